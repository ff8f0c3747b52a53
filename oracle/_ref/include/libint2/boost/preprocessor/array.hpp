# /* **************************************************************************
#  *                                                                          *
#  *     (C) Copyright Paul Mensonides 2002-2011.                             *
#  *     (C) Copyright Edward Diener 2011.                                    *
#  *     Distributed under the Boost Software License, Version 1.0. (See      *
#  *     accompanying file LICENSE_1_0.txt or copy at                         *
#  *     http://www.boost.org/LICENSE_1_0.txt)                                *
#  *                                                                          *
#  ************************************************************************** */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_ARRAY_HPP
# define BOOST_PREPROCESSOR_ARRAY_HPP
#
# include <libint2/boost/preprocessor/array/data.hpp>
# include <libint2/boost/preprocessor/array/elem.hpp>
# include <libint2/boost/preprocessor/array/enum.hpp>
# include <libint2/boost/preprocessor/array/insert.hpp>
# include <libint2/boost/preprocessor/array/pop_back.hpp>
# include <libint2/boost/preprocessor/array/pop_front.hpp>
# include <libint2/boost/preprocessor/array/push_back.hpp>
# include <libint2/boost/preprocessor/array/push_front.hpp>
# include <libint2/boost/preprocessor/array/remove.hpp>
# include <libint2/boost/preprocessor/array/replace.hpp>
# include <libint2/boost/preprocessor/array/reverse.hpp>
# include <libint2/boost/preprocessor/array/size.hpp>
# include <libint2/boost/preprocessor/array/to_list.hpp>
# include <libint2/boost/preprocessor/array/to_seq.hpp>
# include <libint2/boost/preprocessor/array/to_tuple.hpp>
#
# endif
